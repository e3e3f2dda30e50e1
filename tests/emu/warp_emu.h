// warp_emu.h -- runs CUDA kernels written in the warp-synchronous style on the
// host, one fiber per thread.  TEST INFRASTRUCTURE ONLY (there is no GPU in the
// development container; this lets the CPU test-suite execute the device
// SOURCE -- same templates, same indexing, same shared-memory layout -- against
// the oracle).
//
// A CTA is blockDim.x ucontext fibers.  The scheduler steps one warp at a time,
// round-robin over its lanes; every warp-level synchronisation point
// (__syncwarp, a shuffle, a vote, a redux) yields to the next lane, so when a
// lane resumes every other running lane of its warp has reached the same
// point: a yield IS the warp barrier.  Shuffles and votes exchange values
// through a slot per thread: write own slot, yield, read the source slot(s),
// yield.  __syncthreads parks the thread until every thread of the CTA is
// parked or finished.  Warps run one after the other between CTA barriers
// (a legal schedule: the kernels must not depend on inter-warp timing), and
// CTAs run one after the other on a fresh, NaN-filled shared-memory image
// (reads of never-written shared memory poison the result instead of passing
// by luck).  Bulk-TMA copies complete at issue time and mbarrier waits are
// no-ops (qpmpc_b200/csrc/mpc_common.cuh, QPMPC_HOST_EMU).
//
// Checks it makes besides the results: every warp-level synchronisation point
// must be entered by all the lanes it names (divergent_collectives(): on the
// device such a point deadlocks -- this is how the round-1 hang of the
// interior-point kernel, a shuffle inside a short-circuited `&&`, was found);
// running the lanes in descending order as well exposes reads of another
// lane's shared-memory write that have no synchronisation point in between.
// What it cannot show: memory-ordering bugs, inter-warp races, performance.
#pragma once

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <functional>
#include <limits>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__
#define __align__(x)

struct double2 {
    double x, y;
};
struct float2 {
    float x, y;
};

namespace emu {

constexpr int WARP = 32;
enum State { RUN, PARKED, DONE };

struct Idx {
    unsigned x, y, z;
};

struct Cta {
    ucontext_t main_ctx;
    std::vector<ucontext_t> ctx;
    std::vector<char *> stack;  // from a pool that lives as long as the process (never zero-filled)
    std::vector<State> state;
    std::vector<uint64_t> slot;
    std::vector<uint64_t> nsync;     // warp-level synchronisation points entered, per thread
    std::vector<int> site;           // source line of the last one entered
    int nthreads = 0, cur = 0;
    Idx block_idx{0, 0, 0}, block_dim{1, 1, 1}, grid_dim{1, 1, 1};
    std::function<void()> body;
};

// Lane order inside a scheduling round (and warp order inside a CTA round):
// 0 ascending, 1 descending.  A result that depends on it reveals a read of
// another thread's shared-memory write with no synchronisation point in
// between (a data race on the device).
inline int &lane_order() {
    static int order = 0;
    return order;
}

// Collectives entered by only part of the lanes they name (e.g. a shuffle inside
// a short-circuited `a && reduce(b)`, or behind a lane-dependent branch): the
// device deadlocks or returns garbage there, the emulator counts them.
inline long &divergent_collectives() {
    static long count = 0;
    return count;
}

inline Cta *&current() {
    static thread_local Cta *c = nullptr;
    return c;
}
inline void yield() {
    Cta *c = current();
    swapcontext(&c->ctx[c->cur], &c->main_ctx);
}
// Called by every warp-level synchronisation point after its first yield.  The
// lanes scheduled after this one have not run since: each of them that the
// mask names must be waiting at this very point -- same source line, same
// number of points entered.  (Lanes scheduled before may already have left.)
inline void check_convergence(unsigned mask) {
    Cta *c = current();
    const int base = c->cur / WARP * WARP, me = c->cur - base;
    for (int i = 0; i < WARP && base + i < c->nthreads; ++i) {
        const bool later = lane_order() ? i < me : i > me;
        if (!later || !((mask >> i) & 1u) || c->state[base + i] != RUN) continue;
        // __syncwarp (negative line) pairs with a __syncwarp anywhere: bar.warp.sync matches across code
        // addresses; shuffles, votes and reductions exchange data and must be the same instruction
        const bool same_site = c->site[base + i] == c->site[c->cur] || (c->site[base + i] < 0 && c->site[c->cur] < 0);
        if (c->nsync[base + i] != c->nsync[c->cur] || !same_site) {
            if (getenv("EMU_DEBUG") && divergent_collectives() < 12)
                fprintf(stderr, "divergent: lane %d (point %lu, line %d) vs lane %d (point %lu, line %d)\n", me,
                        (unsigned long)c->nsync[c->cur], c->site[c->cur], i, (unsigned long)c->nsync[base + i],
                        c->site[base + i]);
            ++divergent_collectives();
            return;
        }
    }
}
// warp-level synchronisation points entered, summed over all threads (a proxy
// for the length of the dependent chain: tools/emu_sync_profile.py)
inline long &sync_points() {
    static long count = 0;
    return count;
}
// the same, per source line (|line|: __syncwarp is recorded negative), for tools/emu_sync_profile.py --lines
inline std::vector<long> &sync_points_by_line() {
    static std::vector<long> hist(4096, 0);
    return hist;
}
inline void enter_sync(int line) {
    Cta *c = current();
    ++sync_points();
    const int key = line < 0 ? -line : line;
    if (key < (int)sync_points_by_line().size()) ++sync_points_by_line()[key];
    ++c->nsync[c->cur];
    c->site[c->cur] = line;
}
inline void trampoline() {
    Cta *c = current();
    const int t = c->cur;
    c->body();
    c->state[t] = DONE;
    swapcontext(&c->ctx[t], &c->main_ctx);
}
inline int tid() { return current()->cur; }
inline int lane_id() { return current()->cur % WARP; }
inline int warp_base() { return current()->cur / WARP * WARP; }

// Run one CTA of `nthreads` threads to completion.
inline void run_cta(int nthreads, Idx block_idx, Idx grid_dim, const std::function<void()> &body,
                    size_t stack_bytes = 256 << 10) {
    Cta c;
    c.nthreads = nthreads;
    c.block_idx = block_idx;
    c.block_dim = Idx{(unsigned)nthreads, 1, 1};
    c.grid_dim = grid_dim;
    c.body = body;
    c.ctx.resize(nthreads);
    c.stack.resize(nthreads);
    c.state.assign(nthreads, RUN);
    c.slot.assign(nthreads, 0);
    c.nsync.assign(nthreads, 0);
    c.site.assign(nthreads, 0);
    current() = &c;
    static std::vector<char *> pool;
    static size_t pool_bytes = 0;
    if (pool_bytes != stack_bytes) {
        for (char *b : pool) free(b);
        pool.clear();
        pool_bytes = stack_bytes;
    }
    while ((int)pool.size() < nthreads) pool.push_back(static_cast<char *>(malloc(stack_bytes)));
    for (int t = 0; t < nthreads; ++t) {
        c.stack[t] = pool[t];
        getcontext(&c.ctx[t]);
        c.ctx[t].uc_stack.ss_sp = c.stack[t];
        c.ctx[t].uc_stack.ss_size = stack_bytes;
        c.ctx[t].uc_link = &c.main_ctx;
        makecontext(&c.ctx[t], trampoline, 0);
    }
    const int nwarps = (nthreads + WARP - 1) / WARP;
    while (true) {
        for (int wi = 0; wi < nwarps; ++wi) {
            const int w = lane_order() ? nwarps - 1 - wi : wi;  // warps too: a missing __syncthreads shows the same way
            bool any = true;
            while (any) {  // step this warp until every lane is parked or done
                any = false;
                const int lo = w * WARP, hi = std::min(nthreads, (w + 1) * WARP);
                for (int k = lo; k < hi; ++k) {
                    const int t = lane_order() ? hi - 1 - (k - lo) : k;
                    if (c.state[t] != RUN) continue;
                    c.cur = t;
                    swapcontext(&c.main_ctx, &c.ctx[t]);
                    any = any || c.state[t] == RUN;
                }
            }
        }
        bool parked = false;
        for (int t = 0; t < nthreads; ++t) parked = parked || c.state[t] == PARKED;
        if (!parked) break;  // all done
        for (int t = 0; t < nthreads; ++t)
            if (c.state[t] == PARKED) c.state[t] = RUN;  // the CTA barrier opens
    }
    current() = nullptr;
}

template <typename V>
inline void put(V v) {
    static_assert(sizeof(V) <= 8, "slot is 8 bytes");
    Cta *c = current();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(V));
    c->slot[c->cur] = raw;
}
template <typename V>
inline V get(int lane) {
    V out;
    memcpy(&out, &current()->slot[warp_base() + lane], sizeof(V));
    return out;
}
template <typename V>
inline V exchange(int line, V v, int src_lane, unsigned mask = 0xffffffffu) {
    enter_sync(line);
    put(v);
    yield();
    check_convergence(mask);
    const V out = get<V>(src_lane);
    yield();
    return out;
}
// fold the values of the lanes in `mask` (every lane of the mask calls)
template <typename V, typename F>
inline V fold(int line, unsigned mask, V v, F f) {
    enter_sync(line);
    put(v);
    yield();
    check_convergence(mask);
    bool first = true;
    V acc = v;
    for (int i = 0; i < WARP; ++i) {
        if (!((mask >> i) & 1u)) continue;
        const V o = get<V>(i);
        acc = first ? o : f(acc, o);
        first = false;
    }
    yield();
    return acc;
}

}  // namespace emu

// ---- the CUDA surface the emulated sources use -----------------------------
#define threadIdx (emu::Idx{(unsigned)emu::tid(), 0, 0})
#define blockIdx (emu::current()->block_idx)
#define blockDim (emu::current()->block_dim)
#define gridDim (emu::current()->grid_dim)

using std::max;
using std::min;

// The warp-level primitives are macros over emu_*(__LINE__, ...): the source
// line of the call identifies the synchronisation point (check_convergence).
inline void emu_syncwarp(int line, unsigned mask = 0xffffffffu) {
    emu::enter_sync(-line);
    emu::yield();
    emu::check_convergence(mask);
}
inline void __syncthreads() {
    emu::Cta *c = emu::current();
    c->state[c->cur] = emu::PARKED;
    emu::yield();
}
// __syncthreads_or: park, combine, park again (nobody may overwrite its slot
// before everybody has read it).  Every thread of the CTA must call it.
inline int __syncthreads_or(int pred) {
    emu::Cta *c = emu::current();
    emu::put<int>(pred != 0);
    __syncthreads();
    int any = 0;
    for (int t = 0; t < c->nthreads; ++t) {
        int v;
        memcpy(&v, &c->slot[t], sizeof(int));
        any |= v;
    }
    __syncthreads();
    return any;
}
template <typename V>
inline V atomicAdd(V *addr, V v) {  // the schedule is sequential
    const V old = *addr;
    *addr = old + v;
    return old;
}
template <typename V>
inline V emu_shfl_sync(int line, unsigned m, V v, int src, int width = 32) {
    const int lane = emu::lane_id();
    return emu::exchange(line, v, (lane & ~(width - 1)) | (src & (width - 1)), m);
}
template <typename V>
inline V emu_shfl_xor_sync(int line, unsigned m, V v, int mask, int width = 32) {
    const int lane = emu::lane_id(), src = lane ^ mask;
    return emu::exchange(line, v, (src / width == lane / width) ? src : lane, m);
}
template <typename V>
inline V emu_shfl_down_sync(int line, unsigned m, V v, unsigned delta, int width = 32) {
    const int lane = emu::lane_id(), src = lane + (int)delta;
    return emu::exchange(line, v, (src / width == lane / width) ? src : lane, m);
}
inline unsigned emu_ballot_sync(int line, unsigned mask, bool pred) {
    emu::enter_sync(line);
    emu::put<unsigned>(pred ? 1u : 0u);
    emu::yield();
    emu::check_convergence(mask);
    unsigned out = 0;
    for (int i = 0; i < emu::WARP; ++i)
        if (((mask >> i) & 1u) && emu::get<unsigned>(i)) out |= 1u << i;
    emu::yield();
    return out;
}
inline bool emu_all_sync(int line, unsigned m, bool pred) { return emu_ballot_sync(line, m, pred) == m; }
inline bool emu_any_sync(int line, unsigned m, bool pred) { return emu_ballot_sync(line, m, pred) != 0u; }
inline unsigned emu_reduce_max_sync(int line, unsigned m, unsigned v) {
    return emu::fold(line, m, v, [](unsigned a, unsigned b) { return a > b ? a : b; });
}
inline unsigned emu_reduce_min_sync(int line, unsigned m, unsigned v) {
    return emu::fold(line, m, v, [](unsigned a, unsigned b) { return a < b ? a : b; });
}
inline int emu_reduce_max_sync(int line, unsigned m, int v) {
    return emu::fold(line, m, v, [](int a, int b) { return a > b ? a : b; });
}
inline int emu_reduce_min_sync(int line, unsigned m, int v) {
    return emu::fold(line, m, v, [](int a, int b) { return a < b ? a : b; });
}
struct EmuSyncWarpAt {  // __syncwarp() and __syncwarp(mask) through one macro without GNU extensions
    int line;
    void operator()(unsigned mask = 0xffffffffu) const { emu_syncwarp(line, mask); }
};
#define __syncwarp(...) EmuSyncWarpAt{__LINE__}(__VA_ARGS__)
#define __shfl_sync(...) emu_shfl_sync(__LINE__, __VA_ARGS__)
#define __shfl_xor_sync(...) emu_shfl_xor_sync(__LINE__, __VA_ARGS__)
#define __shfl_down_sync(...) emu_shfl_down_sync(__LINE__, __VA_ARGS__)
#define __ballot_sync(...) emu_ballot_sync(__LINE__, __VA_ARGS__)
#define __all_sync(...) emu_all_sync(__LINE__, __VA_ARGS__)
#define __any_sync(...) emu_any_sync(__LINE__, __VA_ARGS__)
#define __reduce_max_sync(...) emu_reduce_max_sync(__LINE__, __VA_ARGS__)
#define __reduce_min_sync(...) emu_reduce_min_sync(__LINE__, __VA_ARGS__)
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __ffs(int v) { return __ffs((unsigned)v); }

inline int __double2hiint(double v) {
    uint64_t r;
    memcpy(&r, &v, 8);
    return (int)(r >> 32);
}
inline int __double2loint(double v) {
    uint64_t r;
    memcpy(&r, &v, 8);
    return (int)(r & 0xffffffffu);
}
inline double __hiloint2double(int hi, int lo) {
    const uint64_t r = ((uint64_t)(unsigned)hi << 32) | (unsigned)lo;
    double v;
    memcpy(&v, &r, 8);
    return v;
}
inline double __longlong_as_double(long long r) {
    double v;
    memcpy(&v, &r, 8);
    return v;
}
inline long long __double_as_longlong(double v) {
    long long r;
    memcpy(&r, &v, 8);
    return r;
}
inline int __float_as_int(float v) {
    int r;
    memcpy(&r, &v, 4);
    return r;
}
inline float __int_as_float(int r) {
    float v;
    memcpy(&v, &r, 4);
    return v;
}
inline float __uint_as_float(unsigned r) {
    float v;
    memcpy(&v, &r, 4);
    return v;
}
inline unsigned __float_as_uint(float v) {
    unsigned r;
    memcpy(&r, &v, 4);
    return r;
}
inline float __frcp_rn(float v) { return 1.0f / v; }
inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
inline double rsqrt(double v) { return 1.0 / sqrt(v); }
