// warp_emu.h — a 32-lane warp on the CPU for the host harnesses under tests/.
//
// Each lane is a fiber (its own stack) running the same callable; __any_sync / __ballot_sync / __shfl_sync / __syncwarp
// suspend the lane until every lane of the warp has arrived at a collective, then all resume with the combined
// result.  Between collectives the lanes run one after the other, which is a legal schedule of independent threads.
// The emulation also does what compute-sanitizer's synccheck does for these call sites: it is an error if the lanes
// wait at DIFFERENT collectives (source lines), or if a lane has returned while others still wait at a full-mask
// collective (on the GPU that is undefined behaviour / a hang).  Only full masks are supported.
#pragma once
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace warp_emu {

constexpr int W = 32;
constexpr size_t STACK = 256 * 1024;

// Context switch between the scheduler and a lane.  glibc's swapcontext makes a sigprocmask system call per switch
// (tens of thousands per emulated env), so on x86-64 a minimal switch of the callee-saved registers and the stack
// pointer is used instead; other hosts fall back to ucontext.
#if defined(__x86_64__)
extern "C" void warp_emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl warp_emu_switch
    .type warp_emu_switch,@function
warp_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size warp_emu_switch,.-warp_emu_switch
)");
typedef void *Context;
#else
typedef ucontext_t Context;
#endif

struct Warp {
    enum State { READY, WAITING, DONE };
    Context main_ctx, lane_ctx[W];
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
    // optional cost model: a lane reports units of work of kind k (work()); between two collectives the warp pays the
    // maximum over its lanes (lock-step execution of divergent lanes), summed into warp_work[k]
    unsigned long long lane_work[W][2] = {}, warp_work[2] = {0, 0};
    void settle() {
        for (int k = 0; k < 2; ++k) {
            unsigned long long m = 0;
            for (int l = 0; l < W; ++l) { if (lane_work[l][k] > m) m = lane_work[l][k]; lane_work[l][k] = 0; }
            warp_work[k] += m;
        }
    }
};

inline unsigned long long (&total_work())[2] {
    static thread_local unsigned long long t[2] = {0, 0};
    return t;
}

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}
inline void work(int kind, int units) {
    Warp *w = current();
    if (w) w->lane_work[w->cur][kind] += (unsigned long long)units;
}

inline void to_lane(Warp *w, int l) {
#if defined(__x86_64__)
    warp_emu_switch(&w->main_ctx, w->lane_ctx[l]);
#else
    swapcontext(&w->main_ctx, &w->lane_ctx[l]);
#endif
}
inline void to_scheduler(Warp *w, int l) {
#if defined(__x86_64__)
    warp_emu_switch(&w->lane_ctx[l], w->main_ctx);
#else
    swapcontext(&w->lane_ctx[l], &w->main_ctx);
#endif
}

inline void lane_entry() {
    Warp *w = current();
    const int l = w->cur;
    w->body(l);
    w->st[l] = Warp::DONE;
    to_scheduler(w, l);  // never resumed
    __builtin_trap();
}

// Run `body(lane)` on 32 lanes.  Returns nullptr, or a description of the convergence error.
inline const char *run(const std::function<void(int)> &body, unsigned long long *n_collectives = nullptr) {
    static thread_local std::vector<char> stacks(W * STACK);  // reused between runs
    Warp w;
    w.body = body;
    current() = &w;
    for (int l = 0; l < W; ++l) {
#if defined(__x86_64__)
        // fresh stack: six zeroed callee-saved registers, then lane_entry as the return address; after the `ret` the
        // stack pointer is 8 below a 16-byte boundary, as at any function entry
        uintptr_t top = reinterpret_cast<uintptr_t>(stacks.data() + (size_t)(l + 1) * STACK) & ~uintptr_t(15);
        void **sp = reinterpret_cast<void **>(top - 64);
        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
        sp[6] = reinterpret_cast<void *>(&lane_entry);
        sp[7] = nullptr;
        w.lane_ctx[l] = sp;
#else
        getcontext(&w.lane_ctx[l]);
        w.lane_ctx[l].uc_stack.ss_sp = stacks.data() + (size_t)l * STACK;
        w.lane_ctx[l].uc_stack.ss_size = STACK;
        w.lane_ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.lane_ctx[l], lane_entry, 0);
#endif
        w.st[l] = Warp::READY;
    }
    for (;;) {
        for (int l = 0; l < W; ++l)
            if (w.st[l] == Warp::READY) {
                w.cur = l;
                to_lane(&w, l);
            }
        int waiting = 0, done = 0, first = -1;
        for (int l = 0; l < W; ++l) {
            if (w.st[l] == Warp::WAITING) { ++waiting; if (first < 0) first = l; }
            else if (w.st[l] == Warp::DONE) ++done;
        }
        w.settle();
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
    total_work()[0] += w.warp_work[0]; total_work()[1] += w.warp_work[1];
    if (n_collectives) *n_collectives = w.collectives;
    return w.error;
}

__attribute__((noinline)) inline uint64_t collective(int op, uint64_t v, int src, const void *site) {
    Warp *w = current();
    const int l = w->cur;
    w->op[l] = op; w->val[l] = v; w->src[l] = src; w->site[l] = site;
    w->st[l] = Warp::WAITING;
    to_scheduler(w, l);
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
