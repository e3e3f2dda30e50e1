// mpc_hessian_tc.cuh -- the Psi' Q Psi contraction of the condensed Hessian on the 5th-generation
// tensor cores (tcgen05 + TMEM), for single-precision problems whose horizon fills a tile.
//
//   P = w_u I + w_x Psi' Psi + w_t psi_N' psi_N                       (qpmpc/mpc_qp.py:99-105)
//
// with Psi in R^(N nx  x  n).  For 32 < n <= 64 and a stage cost this is a (64 x K)(K x 64) GEMM
// per instance, K = (N + 1) nx (260 at N = 64, nx = 4): one CTA per instance stacks
// Xi = [sqrt(w_x) Psi; sqrt(w_t) psi_N] and issues tcgen05.mma.kind::tf32 with the accumulators
// in tensor memory.  TF32 keeps 10 mantissa bits, so every operand is split Xi = hi + lo (hi = the
// value with its low 13 bits cleared, lo = the rest) and the products hi'hi, hi'lo, lo'hi, lo'lo
// are all accumulated in fp32 -- the M = 128 rows of one MMA hold [hi'; lo'] (no padding wasted),
// two MMAs per 8 values of K (B = hi, B = lo) land in two 64-column accumulators:
//     D1 = [hi'; lo'] hi,   D2 = [hi'; lo'] lo,      P = D1[0:64] + D1[64:128] + D2[0:64] + D2[64:128].
// Operands sit in shared memory in the canonical no-swizzle K-major layout (core matrices of 8
// rows x 16 bytes; a probe on the device showed the MN-major no-swizzle form reads as zeros, so
// the operands are staged transposed); one elected thread issues the MMAs, tcgen05.commit signals an mbarrier, the four
// warps read their TMEM lane quadrants with tcgen05.ld for the epilogue.
//
// Device-only (no host-emulation twin: the emulator has no tensor memory); parity is tested on
// the device against the SIMT accumulation and the fp64 oracle (tests/test_gpu_parity.py).
#pragma once

#include "mpc_common.cuh"

#ifndef QPMPC_HOST_EMU
namespace qpmpc {

struct HessianTcParams {
    int batch, N, nx, n;       // n <= 64
    int has_wt, has_wx;
    float w_t, w_x, w_u;
    const float *Psi;          // [batch, N*nx, n]
    const float *psi_last;     // [batch, nx, n]
    float *P;                  // [batch, n, n]
};

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte
    // offset [32,46) (all without their 4 LSBs), version 1 at [46,48), layout type 0 = no swizzle
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// One CTA of 128 threads per instance.  Dynamic shared memory: KB * 4096 bytes of operands
// (per block of 8 K-values: 16 core matrices of hi, 16 of lo), reused by the epilogue.
__global__ void __launch_bounds__(128) mpc_hessian_tc_kernel(const HessianTcParams p) {  // @phase tcgen05 Hessian
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ops = reinterpret_cast<float *>(smem_raw);
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long inst = blockIdx.x;
    const int n = p.n, nx = p.nx;
    const int K = (p.has_wx ? p.N * nx : 0) + (p.has_wt ? nx : 0);
    const int KB = (K + 7) / 8;

    // ---- tensor memory: 128 columns (two 64-column fp32 accumulators), allocated by warp 0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        fence_barrier_init();
    }
    // ---- operands: Xi rows scaled by sqrt(w), split hi / lo, canonical K-major core matrices of
    // A = [hi'; lo'] (128 x K): element (m, k) of block kb = k / 8 sits at
    //     kb*1024 + (j/4)*512 + (m/8)*32 + (m%8)*4 + j%4   floats,  j = k % 8, m = c (hi) or 64 + c (lo)
    // (two core matrices along K, 2048 B apart; groups of 8 rows 128 B apart)
    const float sx = sqrtf(p.w_x), st = sqrtf(p.w_t);
    const float *Psi = p.Psi + (size_t)inst * p.N * nx * n;
    const float *psiN = p.psi_last + (size_t)inst * nx * n;
    const int Kx = p.has_wx ? p.N * nx : 0;
    for (int idx = tid; idx < KB * 8 * 64; idx += 128) {
        const int k = idx >> 6, c = idx & 63;
        float v = 0.f;
        if (c < n && k < K) v = (k < Kx) ? sx * Psi[(size_t)k * n + c] : st * psiN[(size_t)(k - Kx) * n + c];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const float lo = v - hi;
        const int j = k & 7;
        const int off = (k >> 3) * 1024 + (j >> 2) * 512 + (c >> 3) * 32 + (c & 7) * 4 + (j & 3);
        ops[off] = hi;
        ops[off + 256] = lo;  // rows 64 + c: eight groups of 8 rows further
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;

    // ---- MMAs: one elected thread.  A = [hi'; lo'] (128 x 8, K-major), B = hi or lo as a 64 x 8
    // K-major operand: the first / second 64 rows of the same block.  Instruction descriptor
    // (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 64, M = 128.
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t base = smem_u32(ops);
        for (int kb = 0; kb < KB; ++kb) {
            const uint32_t blk = base + (uint32_t)kb * 4096u;
            const uint64_t da = umma_smem_desc(blk, 2048u, 128u);   // LBO: the two K core matrices; SBO: 8-row groups
            const uint64_t dbh = umma_smem_desc(blk, 2048u, 128u);
            const uint64_t dbl = umma_smem_desc(blk + 1024u, 2048u, 128u);
            const uint32_t acc = kb > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                "l"(da), "l"(dbh), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                : "memory");
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem + 64u),
                "l"(da), "l"(dbl), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                : "memory");
        }
        // completion of everything issued so far -> mbarrier (implies tcgen05.fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&mma_bar))
                     : "memory");
    }
    mbar_wait(&mma_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = rows of [hi'; lo'](hi | lo)
    const int row = 32 * warp + lane;
    float acc[64];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r[64];
        const uint32_t addr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(64 * half);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                "%15}, [%16];"
                : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]),
                  "=r"(r[16 * q + 4]), "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]),
                  "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]), "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]),
                  "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]), "=r"(r[16 * q + 14]), "=r"(r[16 * q + 15])
                : "r"(addr + 16u * q));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = (half == 0 ? 0.f : acc[c]) + __uint_as_float(r[c]);
    }
    // rows 64..127 (lo' hi + lo' lo) are added to rows 0..63 through shared memory (operands are dead)
    __syncthreads();
    float *fold = ops;  // [64][65]
    if (row >= 64) {
#pragma unroll
        for (int c = 0; c < 64; ++c) fold[(row - 64) * 65 + c] = acc[c];
    }
    __syncthreads();
    if (row < 64 && row < n) {
        float *Pr = p.P + ((size_t)inst * n + row) * n;
#pragma unroll
        for (int c = 0; c < 64; ++c)
            if (c < n) Pr[c] = acc[c] + fold[row * 65 + c] + (c == row ? p.w_u : 0.f);
    }
    // ---- release tensor memory
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

inline size_t hessian_tc_smem_bytes(int N, int nx, bool has_wt, bool has_wx) {
    const int K = (has_wx ? N * nx : 0) + (has_wt ? nx : 0);
    const size_t ops = (size_t)((K + 7) / 8) * 4096;
    const size_t fold = 64 * 65 * 4;
    return ops > fold ? ops : fold;
}

}  // namespace qpmpc
#endif  // QPMPC_HOST_EMU
