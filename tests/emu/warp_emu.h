// warp_emu.h -- runs warp-synchronous CUDA device code on the host, one fiber
// per lane.  TEST INFRASTRUCTURE ONLY (no GPU in the development container).
//
// A warp is 32 ucontext fibers scheduled round-robin; every warp-level
// synchronisation point (__syncwarp, a shuffle, a vote) yields to the next
// lane, so that when lane 0 resumes every other lane has reached the same
// point: a yield IS the barrier.  Shuffles and votes exchange values through a
// slot array: write own slot, yield, read the source slot(s), yield.  Code
// whose lanes do not reach the same sequence of synchronisation points
// (divergent barriers) misbehaves here just as it is undefined on the device.
//
// Only what the emulated sources use is provided: the *_sync primitives with a
// full mask, the math helpers of mpc_common.cuh, and the qualifier macros.
#pragma once

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline

namespace emu {

constexpr int WARP = 32;

struct Warp {
    ucontext_t main_ctx;
    ucontext_t ctx[WARP];
    std::vector<char> stack[WARP];
    bool finished[WARP];
    int cur = 0;
    uint64_t slot[WARP];
    std::function<void(int)> body;
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

inline void yield() {
    Warp *w = current();
    swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

inline void trampoline() {
    Warp *w = current();
    const int lane = w->cur;
    w->body(lane);
    w->finished[lane] = true;
    swapcontext(&w->ctx[lane], &w->main_ctx);
}

// Run body(lane) for the 32 lanes of one warp to completion.
inline void run_warp(const std::function<void(int)> &body, size_t stack_bytes = 1 << 20) {
    Warp w;
    w.body = body;
    current() = &w;
    for (int i = 0; i < WARP; ++i) {
        w.stack[i].resize(stack_bytes);
        w.finished[i] = false;
        getcontext(&w.ctx[i]);
        w.ctx[i].uc_stack.ss_sp = w.stack[i].data();
        w.ctx[i].uc_stack.ss_size = stack_bytes;
        w.ctx[i].uc_link = &w.main_ctx;
        makecontext(&w.ctx[i], trampoline, 0);
    }
    bool any = true;
    while (any) {
        any = false;
        for (int i = 0; i < WARP; ++i) {
            if (w.finished[i]) continue;
            w.cur = i;
            swapcontext(&w.main_ctx, &w.ctx[i]);
            any = any || !w.finished[i];
        }
    }
    current() = nullptr;
}

inline int lane_id() { return current()->cur; }

template <typename V>
inline V exchange(V v, int src_lane) {
    static_assert(sizeof(V) <= 8, "slot is 8 bytes");
    Warp *w = current();
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(V));
    w->slot[w->cur] = raw;
    yield();
    V out;
    std::memcpy(&out, &w->slot[src_lane], sizeof(V));
    yield();
    return out;
}

}  // namespace emu

// ---- the CUDA surface the emulated sources use -----------------------------
namespace qpmpc {

constexpr unsigned FULL_MASK = 0xffffffffu;

inline void __syncwarp(unsigned = FULL_MASK) { emu::yield(); }

template <typename V>
inline V __shfl_sync(unsigned, V v, int src, int width = 32) {
    const int lane = emu::lane_id();
    return emu::exchange(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <typename V>
inline V __shfl_xor_sync(unsigned, V v, int mask, int width = 32) {
    const int lane = emu::lane_id();
    const int src = lane ^ mask;
    // a source outside the lane's width-segment returns the lane's own value
    return emu::exchange(v, (src / width == lane / width) ? src : lane);
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    emu::Warp *w = emu::current();
    w->slot[w->cur] = pred ? 1u : 0u;
    emu::yield();
    unsigned out = 0;
    for (int i = 0; i < emu::WARP; ++i) out |= (w->slot[i] ? 1u : 0u) << i;
    emu::yield();
    return out;
}
inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }

inline double abs_(double v) { return std::fabs(v); }
inline float abs_(float v) { return std::fabs(v); }
inline double frsqrt_(double v) { return 1.0 / std::sqrt(v); }
inline float frsqrt_(float v) { return 1.0f / std::sqrt(v); }
using std::fmax;
using std::fmin;

template <typename T> struct Num {
    static T inf() { return std::numeric_limits<T>::infinity(); }
    static T nan() { return std::numeric_limits<T>::quiet_NaN(); }
};

}  // namespace qpmpc
